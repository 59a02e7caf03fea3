// D-infinity flow direction stencil: setPosDirDinf + SET2 + VSLOPE fused
// (reference src/dinf.cpp:530-595, 317-373, 286-313).
//
// Eligibility is the D8 rule (global edge / nodata / nodata neighbour -> angle nodata = -FLT_MAX, slope nodata = -1).
// Otherwise the first of the eight triangular facets with the strictly largest VSLOPE slope wins;
// angle = (float)(ANGC*PI/2 + ANGF*A), slope = (float)Smax; no facet with S > 0 -> angle -1 (flat), slope 0.
//
// Structure (persistent CTAs, 2-D TMA tiles through a three-stage ring, tile_pipe.cuh):
//  1. the eight facets are ranked with a float copy of VSLOPE's three-branch formula on squared slopes (no root);
//     the float error is ~1e-6, so every facet whose float rank is more than 4e-5 below the best one cannot be the
//     exact winner.  In all but ~1e-4 of the cells exactly one facet survives: the winner is known without FP64.
//     Cells with several survivors evaluate those facets in FP64 in increasing K with the reference's strict '>'.
//  2. the winner is evaluated ONCE per cell, uniformly over the warp (no per-facet divergence): both quotients,
//     the clipped slope, the root, then selects; the three divisions are divisions by row constants
//     (div_const, rowfact.cuh: five FMA-pipe FP64 operations each, correctly rounded), atan2 only for the facets whose
//     direction lies inside the facet.  Bit-identical to the reference's FP64 arithmetic.
// HBM traffic per cell: read fel 4 B, write ang 4 B + slp 4 B = 12 B (algorithmic).
#include "dinf_common.cuh"
#include "kernels.h"
#include "rowfact.cuh"
#include "tile_pipe.cuh"

namespace td {
namespace {
constexpr int TW = 128, TH = 32, STAGES = 3;
using Ring = TileRing<float, TW, TH, STAGES>;

// VSLOPE with the divisions by row constants done through their reciprocals (rf.safe) or as plain divisions
__device__ __forceinline__ Facet vslope_row(double E0, double E1, double E2, double D1, double D2, double DD, double r1, double r2,
                                            double rd, bool safe) {
  Facet f;
  const double x1 = E0 - E1, x2 = E1 - E2, x3 = E0 - E2;
  double Sc;
  if (safe) { f.S1 = div_const(x1, D1, r1); f.S2 = div_const(x2, D2, r2); Sc = div_const(x3, DD, rd); }
  else { f.S1 = x1 / D1; f.S2 = x2 / D2; Sc = x3 / DD; }
  if (f.S2 < 0.) { f.S = f.S1; f.code = 0; return f; }
  bool clip;
  if (f.S1 <= 0.) clip = !(f.S1 == 0. && f.S2 == 0.);
  else {
    const double x = f.S2 * D1, y = f.S1 * D2;
    if (x > y * (1. + 1e-9)) clip = true;
    else if (x < y * (1. - 1e-9)) clip = false;
    else clip = atan2(f.S2, f.S1) > atan2(D2, D1);
  }
  const double Sq = sqrt(f.S1 * f.S1 + f.S2 * f.S2);
  if (clip) { f.S = Sc; f.code = 1; return f; }
  f.S = Sq;
  f.code = (f.S1 == 0. && f.S2 == 0.) ? 0 : 2;
  return f;
}

__global__ void __launch_bounds__(256) k_dinf_stencil(const TD_GRID_CONSTANT TileMap tm, float* __restrict__ ang, float* __restrict__ slp,
                                                      const RowFact* __restrict__ rowf, Strip s, float nodata,
                                                      unsigned long long* __restrict__ nflat) {
  extern __shared__ __align__(128) unsigned char dsm128[];
  using G = Ring::G;
  Ring ring;
  ring.init(dsm128, &tm, s);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned myflat = 0;
  for (long long t = blockIdx.x; t < ring.ntiles; t += gridDim.x) {
    int r0, c0;
    const float* tile = ring.acquire(t, r0, c0);
#pragma unroll 1
    for (int pass = 0; pass < TH / 8; ++pass) {
      const int tr = warp * (TH / 8) + pass;
      const int r = r0 + tr, c = c0 + lane * 4;
      if (r > s.ny || c >= s.pitch) continue;
      const float* pm = tile + tr * G::SW + G::HP + lane * 4;
      float nb[3][6];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float* q = pm + j * G::SW;
        const float4 v = *reinterpret_cast<const float4*>(q);
        nb[j][0] = q[-1]; nb[j][1] = v.x; nb[j][2] = v.y; nb[j][3] = v.z; nb[j][4] = v.w; nb[j][5] = q[4];
      }
      const RowFact rf = rowf[r - 1];
      float colmin[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) colmin[i] = fminf(fminf(fabsf(nb[0][i] - nodata), fabsf(nb[1][i] - nodata)), fabsf(nb[2][i] - nodata));
      unsigned em = ((r == 1 && !s.has_top) || (r == s.ny && !s.has_bot)) ? 0xfu : 0u;
      em |= (c == 0) ? 1u : 0u;
      const int klast = s.nx - 1 - c;
      if (klast < 4) em |= (0xfu << max(klast, 0)) & 0xfu;
      float oa[4], os[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float z = nb[1][i + 1];
        const bool bad = (fminf(fminf(colmin[i], colmin[i + 1]), colmin[i + 2]) < TD_MINEPS) || ((em >> i) & 1u);
        // ---- 1. float ranking of the eight facets (squared slopes, 0 for a non-positive one)
        float st[8], smaxf = 0.f;
#pragma unroll
        for (int K = 1; K <= 8; ++K) {
          const float e1 = nb[1 + fI1(K)][i + 1 + fJ1(K)], e2 = nb[1 + fI2(K)][i + 1 + fJ2(K)];
          const float r1 = fD1isDx(K) ? rf.rdxf : rf.rdyf, r2 = fD1isDx(K) ? rf.rdyf : rf.rdxf;
          const float d1f = fD1isDx(K) ? rf.dxf : rf.dyf, d2f = fD1isDx(K) ? rf.dyf : rf.dxf;
          const float s1 = (z - e1) * r1, s2 = (e1 - e2) * r2;
          const bool clip = (s1 <= 0.f) ? !(s1 == 0.f && s2 == 0.f) : (s2 * d1f > s1 * d2f);
          const float lin = (s2 < 0.f) ? s1 : (z - e2) * rf.rddf;
          const float Q = (s2 < 0.f || clip) ? (lin > 0.f ? lin * lin : 0.f) : s1 * s1 + s2 * s2;
          st[K - 1] = Q;
          smaxf = fmaxf(smaxf, Q);
        }
        unsigned cand = 0;
        const float thr = smaxf * 0.99996f;
#pragma unroll
        for (int K = 1; K <= 8; ++K) cand |= (st[K - 1] >= thr && st[K - 1] > 0.f) ? (1u << (K - 1)) : 0u;
        if (bad) cand = 0;
        const float* q0 = pm + G::SW + i;          // centre cell in the staged tile
        const double E0 = (double)z;
        int KD = cand ? __ffs(cand) : 0;
        if (cand & (cand - 1u)) {
          // several facets within the float error of the best one: the exact slopes decide (increasing K, strict '>')
          double SMAX = 0.; KD = 0;
          for (unsigned m = cand; m; m &= m - 1) {
            const int K = __ffs(m);
            const bool d1x = fD1isDx(K);
            const float e1 = q0[fI1(K) * G::SW + fJ1(K)], e2 = q0[fI2(K) * G::SW + fJ2(K)];
            const Facet f = vslope_row(E0, (double)e1, (double)e2, d1x ? rf.dx : rf.dy, d1x ? rf.dy : rf.dx, rf.dd,
                                       d1x ? rf.rdx : rf.rdy, d1x ? rf.rdy : rf.rdx, rf.rdd, rf.safe != 0);
            if (f.S > SMAX) { SMAX = f.S; KD = K; }
          }
        }
        // ---- 2. the winner, once and uniformly
        float a = -1.0f, sl = 0.f;
        {
          const int K = KD ? KD : 1;
          // facet geometry as arithmetic on K (no table): E1 offset, E2 offset, which cell size is D1
          const int i1 = (K == 2 || K == 3) ? -1 : (K == 6 || K == 7) ? 1 : 0;
          const int j1 = (K == 1 || K == 8) ? 1 : (K == 4 || K == 5) ? -1 : 0;
          const int i2 = K <= 4 ? -1 : 1;
          const int j2 = (K == 1 || K == 2 || K == 7 || K == 8) ? 1 : -1;
          const bool d1x = (K == 1 || K == 4 || K == 5 || K == 8);
          const float e1 = q0[i1 * G::SW + j1], e2 = q0[i2 * G::SW + j2];
          const Facet f = vslope_row(E0, (double)e1, (double)e2, d1x ? rf.dx : rf.dy, d1x ? rf.dy : rf.dx, rf.dd,
                                     d1x ? rf.rdx : rf.rdy, d1x ? rf.rdy : rf.rdx, rf.rdd, rf.safe != 0);
          if (KD > 0 && f.S > 0.) {
            const double A = f.code == 0 ? 0. : f.code == 1 ? (d1x ? rf.adA : rf.adB) : atan2(f.S2, f.S1);
            a = dinf_angle(K, A);
            sl = (float)f.S;
          } else KD = 0;
        }
        oa[i] = bad ? TD_MISSINGFLOAT : a;
        os[i] = bad ? -1.0f : sl;
        if (!bad && KD == 0) ++myflat;
      }
      const long long o = s.idx(r, c);
      *reinterpret_cast<float4*>(ang + o) = make_float4(oa[0], oa[1], oa[2], oa[3]);
      *reinterpret_cast<float4*>(slp + o) = make_float4(os[0], os[1], os[2], os[3]);
    }
    ring.release(&tm, t);
  }
  // flat count: warp reduce, then one atomic per CTA
  __shared__ unsigned wflat[8];
  for (int o = 16; o; o >>= 1) myflat += __shfl_xor_sync(0xffffffffu, myflat, o);
  if (lane == 0) wflat[warp] = myflat;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned tt = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) tt += wflat[i];
    if (tt) atomicAdd(nflat, (unsigned long long)tt);
  }
}
}  // namespace

int launch_dinf_stencil(const float* elev, float* ang, float* slp, const RowFact* rowf, const Strip& s, float nodata,
                        unsigned long long* nflat, cudaStream_t st) {
  TileMap tm;
  if (int rc = make_tile_map(&tm, elev, 4, s.pitch, s.ny + 2, Ring::G::SW, Ring::G::ROWS)) return rc;
  const long long ntiles = (long long)((s.pitch + TW - 1) / TW) * ((s.ny + TH - 1) / TH);
  int grid = 0;
  if (int rc = stencil_grid((const void*)k_dinf_stencil, 256, Ring::SMEM, ntiles, &grid)) return rc;
  k_dinf_stencil<<<grid, 256, Ring::SMEM, st>>>(tm, ang, slp, rowf, s, nodata, nflat);
  TD_LAUNCHED();
  TD_CUDA(cudaGetLastError());
  return TD_OK;
}
}  // namespace td
