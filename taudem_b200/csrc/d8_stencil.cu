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

// One cell.  nb = the staged neighbourhood (row above / centre / below, columns i..i+2 of nb).
// The reference scans k = 1,3,5,7,2,4,6,8 and keeps the first k with the strictly largest
// slope_k = (float)(fact_k * (double)(z - z_k)).  fact takes only three values per row (E/W, N/S,
// diagonals) and the rounding is monotone in the elevation drop, so the maximum of each group is
// attained by the group's largest drop: three exact products instead of eight.  Within a group the
// winner is the first member (scan order) whose slope equals the group's: a drop more than 2^-20
// (relative) below the largest one cannot round to the same slope, so the candidate is the first
// member inside that band; if it is not the largest drop itself the cell is ambiguous (two nearly equal
// drops, rare) and takes the literal eight-product path.  A group whose largest drop is <= 0 has no
// member inside the band and can never win (S > 0 is required), so it never raises the flag.
__device__ __forceinline__ bool d8_cell(const float (&nb)[3][6], int i, double fE, double fN, double fD, int& dir, float& smax) {
  const float z = nb[1][i + 1];
  const float e1 = z - nb[1][i + 2], e5 = z - nb[1][i], e3 = z - nb[0][i + 1], e7 = z - nb[2][i + 1];
  const float e2 = z - nb[0][i + 2], e4 = z - nb[0][i], e6 = z - nb[2][i], e8 = z - nb[2][i + 2];
  const float m15 = fmaxf(e1, e5), m37 = fmaxf(e3, e7), mD = fmaxf(fmaxf(e2, e4), fmaxf(e6, e8));
  const float sE = (float)(fE * (double)m15), sN = (float)(fN * (double)m37), sD = (float)(fD * (double)mD);
  const float S = fmaxf(fmaxf(sE, sN), sD);
  const float c = 0.99999905f;                       // 1 - 2^-20
  const float tE = m15 * c, tN = m37 * c, tD = mD * c;
  // candidates coded as (scan position << 4) | k so that an integer minimum picks the earliest one
  const bool in1 = e1 > tE, in3 = e3 > tN, in2 = e2 > tD, in4 = e4 > tD, in6 = e6 > tD;
  int cE = in1 ? 0x01 : 0x25;
  int cN = in3 ? 0x13 : 0x37;
  int cD = in2 ? 0x42 : (in4 ? 0x54 : (in6 ? 0x66 : 0x78));
  const float eD = in2 ? e2 : (in4 ? e4 : (in6 ? e6 : mD));
  const bool amb = (in1 & (e1 < m15)) | (in3 & (e3 < m37)) | (eD < mD);
  cE = (sE == S) ? cE : 0xff; cN = (sN == S) ? cN : 0xff; cD = (sD == S) ? cD : 0xff;
  const int best = min(cE, min(cN, cD));
  const bool pos = S > 0.f;
  dir = pos ? (best & 15) : 0;
  smax = pos ? S : 0.f;
  return amb & pos;
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
    // nodata: distance of every staged value to the nodata value, minimum per column, then per 3x3 window
    float colmin[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) colmin[i] = fminf(fminf(fabsf(nb[0][i] - nodata), fabsf(nb[1][i] - nodata)), fabsf(nb[2][i] - nodata));
    // cells on the edge of the whole grid or beyond the last column, as a 4-bit mask for this thread's cells
    unsigned em = ((r == 1 && !s.has_top) || (r == s.ny && !s.has_bot)) ? 0xfu : 0u;
    em |= (c == 0) ? 1u : 0u;
    const int klast = s.nx - 1 - c;                                   // cell index of the last grid column within this thread
    if (klast < 4) em |= (0xfu << max(klast, 0)) & 0xfu;
    short od[4]; float os[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool bad = (fminf(fminf(colmin[i], colmin[i + 1]), colmin[i + 2]) < TD_MINEPS) || ((em >> i) & 1u);
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
  // flat count: warp reduce, then one atomic per CTA (a million CTAs at 65536^2 all add to the same word)
  __shared__ unsigned wflat[8];
  for (int o = 16; o; o >>= 1) myflat += __shfl_xor_sync(0xffffffffu, myflat, o);
  if (lane == 0) wflat[warp] = myflat;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += wflat[i];
    if (t) atomicAdd(nflat, (unsigned long long)t);
  }
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
