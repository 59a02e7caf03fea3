// D-infinity flow direction stencil: setPosDirDinf + SET2 + VSLOPE fused
// (reference src/dinf.cpp:530-595, 317-373, 286-313).
//
// Eligibility is the D8 rule (global edge / nodata / nodata neighbour -> angle
// nodata = -FLT_MAX, slope nodata = -1).  Otherwise the eight triangular facets are
// evaluated in double precision exactly as VSLOPE does; the first facet with the
// strictly largest slope wins; angle = (float)(ANGC*PI/2 + ANGF*A), slope = (float)Smax;
// no facet with S > 0 -> angle -1 (flat), slope 0.
// The eight facets are first ranked with a float copy of the same formula (squared slopes, no root); only the facets within 1e-3
// of the best approximate slope (1-2 in practice) are evaluated in FP64, in increasing K with the
// reference's strict '>' — the exact winner is always among them.
// HBM traffic per cell: read fel 4 B, write ang 4 B + slp 4 B = 12 B (algorithmic);
// the kernel is FP64-issue bound (4 + 8 divisions, up to 8 square roots per cell).
#include "dinf_common.cuh"

namespace td {
namespace {
constexpr int TW = 128, TH = 32;

__global__ void __launch_bounds__(256) k_dinf_stencil(const float* __restrict__ elev, float* __restrict__ ang,
                                                      float* __restrict__ slp, const double* __restrict__ dxc,
                                                      const double* __restrict__ dyc, const double* __restrict__ thA,
                                                      const double* __restrict__ thB, Strip s, float nodata,
                                                      unsigned long long* __restrict__ nflat) {
  using G = TileGeom<float, TW, TH>;
  __shared__ __align__(128) float tile[G::ELEMS];
  __shared__ __align__(8) uint64_t bar;
  const int c0 = blockIdx.x * TW, r0 = 1 + blockIdx.y * TH;
  load_tile_tma<float, TW, TH>(tile, &bar, elev, s, r0, c0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned myflat = 0;
#pragma unroll 1
  for (int pass = 0; pass < TH / 8; ++pass) {
    const int tr = warp + 8 * pass;
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
    const double dx = dxc[r - 1], dy = dyc[r - 1];
    const double DD = sqrt(dx * dx + dy * dy);
    const double adA = thA[r - 1], adB = thB[r - 1];   // atan2(dy,dx), atan2(dx,dy)
    const float dxf = (float)dx, dyf = (float)dy, rdx = 1.0f / dxf, rdy = 1.0f / dyf, rdd = 1.0f / (float)DD;
    float oa[4], os[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cc = c + i;
      const float z = nb[1][i + 1];
      bool bad = nd_f(z, nodata) || s.global_edge(r, cc) || cc >= s.nx;
#pragma unroll
      for (int k = 1; k <= 8; ++k) bad = bad || nd_f(nb[1 + drow(k)][i + 1 + dcol(k)], nodata);
      // float pre-screen of the eight facets (same three-branch formula as VSLOPE, ~1e-6 relative): only the
      // facets within 1e-3 of the best approximate slope can be the exact winner and get the FP64 treatment
      float st[8], smaxf = 0.f;
#pragma unroll
      for (int K = 1; K <= 8; ++K) {
        const float e1 = nb[1 + fI1(K)][i + 1 + fJ1(K)], e2 = nb[1 + fI2(K)][i + 1 + fJ2(K)];
        const float r1 = fD1isDx(K) ? rdx : rdy, r2 = fD1isDx(K) ? rdy : rdx;
        const float d1f = fD1isDx(K) ? dxf : dyf, d2f = fD1isDx(K) ? dyf : dxf;
        const float s1 = (z - e1) * r1, s2 = (e1 - e2) * r2;
        const bool clip = (s1 <= 0.f) ? !(s1 == 0.f && s2 == 0.f) : (s2 * d1f > s1 * d2f);
        // squared slope (0 for a non-positive one): the ranking needs no square root
        const float lin = (s2 < 0.f) ? s1 : (z - e2) * rdd;
        const float Q = (s2 < 0.f || clip) ? (lin > 0.f ? lin * lin : 0.f) : s1 * s1 + s2 * s2;
        st[K - 1] = Q;
        smaxf = fmaxf(smaxf, Q);
      }
      unsigned cand = 0;
      const float thr = smaxf * 0.998f;            // (0.999)^2
#pragma unroll
      for (int K = 1; K <= 8; ++K) cand |= (st[K - 1] >= thr && st[K - 1] > 0.f) ? (1u << (K - 1)) : 0u;
      if (bad) cand = 0;
      double SMAX = 0.; int KD = 0; Facet best; best.S = 0.; best.S1 = best.S2 = 0.; best.code = 0;
      const float* q0 = pm + G::SW + i;          // centre cell in the staged tile
      const double E0 = (double)z;
      for (unsigned m = cand; m; m &= m - 1) {     // increasing K, strict '>' : first facet with the largest exact slope
        const int K = __ffs(m);
        const bool d1x = fD1isDx(K);
        const float e1 = q0[fI1(K) * G::SW + fJ1(K)], e2 = q0[fI2(K) * G::SW + fJ2(K)];
        const Facet f = vslope_dev(E0, (double)e1, (double)e2, d1x ? dx : dy, d1x ? dy : dx, DD);
        if (f.S > SMAX) { SMAX = f.S; KD = K; best = f; }
      }
      float a = -1.0f;
      if (KD > 0) a = dinf_angle(KD, facet_angle(best, fD1isDx(KD) ? adA : adB));
      oa[i] = bad ? TD_MISSINGFLOAT : a;
      os[i] = bad ? -1.0f : (float)SMAX;
      if (!bad && KD == 0) ++myflat;
    }
    const long long o = s.idx(r, c);
    *reinterpret_cast<float4*>(ang + o) = make_float4(oa[0], oa[1], oa[2], oa[3]);
    *reinterpret_cast<float4*>(slp + o) = make_float4(os[0], os[1], os[2], os[3]);
  }
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

cudaError_t launch_dinf_stencil(const float* elev, float* ang, float* slp, const double* dxc, const double* dyc,
                                const double* thA, const double* thB, const Strip& s, float nodata,
                                unsigned long long* nflat, cudaStream_t st) {
  dim3 grid((s.pitch + TW - 1) / TW, (s.ny + TH - 1) / TH);
  k_dinf_stencil<<<grid, 256, 0, st>>>(elev, ang, slp, dxc, dyc, thA, thB, s, nodata, nflat);
  TD_LAUNCHED();
  return cudaGetLastError();
}
}  // namespace td
