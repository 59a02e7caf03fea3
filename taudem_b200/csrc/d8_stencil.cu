// D8 steepest-descent stencil: setPosDir + setFlow + calcSlope fused
// (reference src/d8.cpp:359-409, 103-150, 153-177).
//
//  * a cell on the global grid edge, a nodata cell, or a cell with any nodata
//    8-neighbour gets dir = -32768 and slope = -1;
//  * otherwise k is scanned in the order 1,3,5,7,2,4,6,8 with a strict '>' on
//    slope_k = (float)(fact[j][k] * (double)(float)(z0 - zk)); dir 0 = flat;
//  * slope = slope of the chosen direction, 0 on flats.
// The reference's dontCross / "neighbour points back" tests cannot fire in this
// pass (a strict steepest descent never crosses or faces another one), so the
// pass is a pure 3x3 stencil.
//
// HBM traffic per cell: read fel 4 B, write p 2 B + sd8 4 B = 10 B (algorithmic).
// Tile: 32 rows x 128 columns per CTA, staged through shared memory by 1-D TMA
// bulk copies (34 row copies of 544 B), each thread produces 4 adjacent cells of
// one row and stores them as one short4 + one float4.
#include "common.cuh"

namespace td {

namespace {
constexpr int TW = 128, TH = 32;

__device__ __forceinline__ void d8_try(float z, float zn, double f, int k, float& smax, int& dir) {
  const float diff = z - zn;
  const float sl = (float)(f * (double)diff);
  if (sl > smax) { smax = sl; dir = k; }
}

// Literal reference order with all eight products (rare path, out of line; q = centre cell in the staged tile)
__device__ __noinline__ void d8_literal(const float* q, int sw, double fE, double fN, double fD, int* dir, float* smax) {
  const float z = q[0];
  int d = 0; float sm = 0.f;
  d8_try(z, q[1], fE, 1, sm, d); d8_try(z, q[-sw], fN, 3, sm, d);
  d8_try(z, q[-1], fE, 5, sm, d); d8_try(z, q[sw], fN, 7, sm, d);
  d8_try(z, q[-sw + 1], fD, 2, sm, d); d8_try(z, q[-sw - 1], fD, 4, sm, d);
  d8_try(z, q[sw - 1], fD, 6, sm, d); d8_try(z, q[sw + 1], fD, 8, sm, d);
  *dir = d; *smax = sm;
}

// One cell.  nbr = the 3x3 neighbourhood (row above / centre / below, columns i..i+2 of nb).
// The reference scans k = 1,3,5,7,2,4,6,8 and keeps the first k with the strictly largest
// slope_k = (float)(fact_k * (double)(z - z_k)).  fact takes only three values per row (E/W, N/S,
// diagonals) and the rounding is monotone in the elevation drop, so the maximum of each group is
// attained by the group's largest drop: three exact products instead of eight.  The winner is the
// first k in scan order whose group slope equals the maximum and whose drop equals the group's
// largest drop.  Two different drops can round to the same slope only when they are within an ulp or
// two of each other; any such near-tie (relative gap < 2^-20) takes the literal eight-product path.
__device__ __forceinline__ bool d8_cell(const float (&nb)[3][6], int i, double fE, double fN, double fD, int& dir, float& smax) {
  const float z = nb[1][i + 1];
  const float e1 = z - nb[1][i + 2], e5 = z - nb[1][i], e3 = z - nb[0][i + 1], e7 = z - nb[2][i + 1];
  const float e2 = z - nb[0][i + 2], e4 = z - nb[0][i], e6 = z - nb[2][i], e8 = z - nb[2][i + 2];
  const float m15 = fmaxf(e1, e5), m37 = fmaxf(e3, e7), m24 = fmaxf(e2, e4), m68 = fmaxf(e6, e8), mD = fmaxf(m24, m68);
  const float sE = (float)(fE * (double)m15), sN = (float)(fN * (double)m37), sD = (float)(fD * (double)mD);
  const float S = fmaxf(fmaxf(sE, sN), sD);
  // candidate of each group = its first member (scan order 1,3,5,7,2,4,6,8) with the largest drop,
  // coded as (scan position << 4) | k so that an integer minimum picks the earliest one
  int cE = (e1 >= e5) ? 0x01 : 0x25;
  int cN = (e3 >= e7) ? 0x13 : 0x37;
  const int c24 = (e2 >= e4) ? 0x42 : 0x54, c68 = (e6 >= e8) ? 0x66 : 0x78;
  int cD = (m24 >= m68) ? c24 : c68;
  cE = (sE == S) ? cE : 0xff; cN = (sN == S) ? cN : 0xff; cD = (sD == S) ? cD : 0xff;
  const int best = min(cE, min(cN, cD));
  const bool pos = S > 0.f;
  dir = pos ? (best & 15) : 0;
  smax = pos ? S : 0.f;
  // a drop within 2^-20 (relative) below its group's largest drop could round to the same slope and, if it
  // comes earlier in the scan, win: such cells (rare) take the literal path.  Bitwise logic: no branches.
  const float c = 0.99999905f;
  const float tE = m15 * c, tN = m37 * c, tD = mD * c;
  const float lE = fminf(e1, e5), lN = fminf(e3, e7);
  const bool near = pos & (((lE < m15) & (lE > tE)) | ((lN < m37) & (lN > tN)) | ((e2 < mD) & (e2 > tD)) | ((e4 < mD) & (e4 > tD)) |
                           ((e6 < mD) & (e6 > tD)) | ((e8 < mD) & (e8 > tD)));
  return near;
}

__global__ void __launch_bounds__(256) k_d8_stencil(const float* __restrict__ elev, short* __restrict__ dir,
                                                    float* __restrict__ slope, const double* __restrict__ dxc,
                                                    const double* __restrict__ dyc, Strip s, float nodata,
                                                    unsigned long long* __restrict__ nflat) {
  using G = TileGeom<float, TW, TH>;
  __shared__ __align__(128) float tile[G::ELEMS];
  __shared__ __align__(8) uint64_t bar;
  __shared__ double sfact[TH][3];               // 1/sqrt((d1 dx)^2 + (d2 dy)^2) per tile row: E/W, N/S, diagonal (src/d8.cpp:369-377)
  const int c0 = blockIdx.x * TW, r0 = 1 + blockIdx.y * TH;
  if (threadIdx.x < TH && r0 + threadIdx.x <= s.ny) {
    const double dx = dxc[r0 + threadIdx.x - 1], dy = dyc[r0 + threadIdx.x - 1];
    sfact[threadIdx.x][0] = 1. / sqrt(dx * dx);
    sfact[threadIdx.x][1] = 1. / sqrt(dy * dy);
    sfact[threadIdx.x][2] = 1. / sqrt(dx * dx + dy * dy);
  }
  load_tile_tma<float, TW, TH>(tile, &bar, elev, s, r0, c0);   // contains the __syncthreads() that publishes sfact

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned myflat = 0;
#pragma unroll 1
  for (int pass = 0; pass < TH / 8; ++pass) {
    const int tr = warp + 8 * pass;
    const int r = r0 + tr, c = c0 + lane * 4;
    if (r > s.ny || c >= s.pitch) continue;
    const float* pm = tile + tr * G::SW + G::HP + lane * 4;   // row above, column c
    float nb[3][6];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float* p = pm + j * G::SW;
      const float4 v = *reinterpret_cast<const float4*>(p);
      nb[j][0] = p[-1]; nb[j][1] = v.x; nb[j][2] = v.y; nb[j][3] = v.z; nb[j][4] = v.w; nb[j][5] = p[4];
    }
    const double fE = sfact[tr][0], fN = sfact[tr][1], fD = sfact[tr][2];
    // nodata per staged value, then per column, then per 3x3 window
    bool colbad[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) colbad[i] = nd_f(nb[0][i], nodata) || nd_f(nb[1][i], nodata) || nd_f(nb[2][i], nodata);
    short od[4]; float os[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cc = c + i;
      const bool bad = colbad[i] || colbad[i + 1] || colbad[i + 2] || s.global_edge(r, cc) || cc >= s.nx;
      int d; float smax;
      if (d8_cell(nb, i, fE, fN, fD, d, smax)) d8_literal(pm + G::SW + i, G::SW, fE, fN, fD, &d, &smax);
      od[i] = bad ? TD_MISSINGSHORT : (short)d;
      os[i] = bad ? -1.0f : smax;
      if (!bad && d == 0) ++myflat;
    }
    const long long o = s.idx(r, c);
    *reinterpret_cast<short4*>(dir + o) = make_short4(od[0], od[1], od[2], od[3]);
    *reinterpret_cast<float4*>(slope + o) = make_float4(os[0], os[1], os[2], os[3]);
  }
  // flat count: warp reduce, one atomic per warp that saw flats
  for (int o = 16; o; o >>= 1) myflat += __shfl_xor_sync(0xffffffffu, myflat, o);
  if (lane == 0 && myflat) atomicAdd(nflat, (unsigned long long)myflat);
}
}  // namespace

cudaError_t launch_d8_stencil(const float* elev, short* dir, float* slope, const double* dxc, const double* dyc,
                              const Strip& s, float nodata, unsigned long long* nflat, cudaStream_t st) {
  dim3 grid((s.pitch + TW - 1) / TW, (s.ny + TH - 1) / TH);
  k_d8_stencil<<<grid, 256, 0, st>>>(elev, dir, slope, dxc, dyc, s, nodata, nflat);
  TD_LAUNCHED();
  return cudaGetLastError();
}

}  // namespace td
