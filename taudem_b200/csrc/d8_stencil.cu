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

__global__ void __launch_bounds__(256) k_d8_stencil(const float* __restrict__ elev, short* __restrict__ dir,
                                                    float* __restrict__ slope, const double* __restrict__ dxc,
                                                    const double* __restrict__ dyc, Strip s, float nodata,
                                                    unsigned long long* __restrict__ nflat) {
  using G = TileGeom<float, TW, TH>;
  __shared__ __align__(128) float tile[G::ELEMS];
  __shared__ __align__(8) uint64_t bar;
  const int c0 = blockIdx.x * TW, r0 = 1 + blockIdx.y * TH;
  load_tile_tma<float, TW, TH>(tile, &bar, elev, s, r0, c0);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned myflat = 0;
#pragma unroll
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
    const double dx = dxc[r - 1], dy = dyc[r - 1];
    const double fE = 1. / sqrt(dx * dx), fN = 1. / sqrt(dy * dy), fD = 1. / sqrt(dx * dx + dy * dy);
    bool ndv[3][6];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int i = 0; i < 6; ++i) ndv[j][i] = nd_f(nb[j][i], nodata);

    short od[4]; float os[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cc = c + i;
      const float z = nb[1][i + 1];
      bool bad = ndv[1][i + 1] || s.global_edge(r, cc) || cc >= s.nx;
      bad = bad || ndv[0][i] || ndv[0][i + 1] || ndv[0][i + 2] || ndv[1][i] || ndv[1][i + 2] || ndv[2][i] ||
            ndv[2][i + 1] || ndv[2][i + 2];
      int d = 0; float smax = 0.f;
      d8_try(z, nb[1][i + 2], fE, 1, smax, d);
      d8_try(z, nb[0][i + 1], fN, 3, smax, d);
      d8_try(z, nb[1][i], fE, 5, smax, d);
      d8_try(z, nb[2][i + 1], fN, 7, smax, d);
      d8_try(z, nb[0][i + 2], fD, 2, smax, d);
      d8_try(z, nb[0][i], fD, 4, smax, d);
      d8_try(z, nb[2][i], fD, 6, smax, d);
      d8_try(z, nb[2][i + 2], fD, 8, smax, d);
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
