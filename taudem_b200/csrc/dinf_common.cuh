// D-infinity facet arithmetic shared by the stencil and the flat-resolution kernels.
// reference: VSLOPE src/dinf.cpp:286-313, facet tables src/dinf.cpp:328-335.
#pragma once
#include <string.h>

#include "common.cuh"

namespace td {

// Facet K = 1..8: E1 is the cardinal neighbour, E2 the diagonal one.
//   (row,col) offsets: I1/J1 for E1, I2/J2 for E2; D1/D2 = dx or dy by ID1/ID2.
__host__ __device__ __forceinline__ constexpr int fI1(int K) { return (K == 2 || K == 3) ? -1 : (K == 6 || K == 7) ? 1 : 0; }
__host__ __device__ __forceinline__ constexpr int fJ1(int K) { return (K == 1 || K == 8) ? 1 : (K == 4 || K == 5) ? -1 : 0; }
__host__ __device__ __forceinline__ constexpr int fI2(int K) { return (K >= 1 && K <= 4) ? -1 : 1; }
__host__ __device__ __forceinline__ constexpr int fJ2(int K) { return (K == 1 || K == 2 || K == 7 || K == 8) ? 1 : -1; }
__host__ __device__ __forceinline__ constexpr bool fD1isDx(int K) { return K == 1 || K == 4 || K == 5 || K == 8; }
__host__ __device__ __forceinline__ constexpr double fANGC(int K) { return (double)(K / 2); }            // 0,1,1,2,2,3,3,4
__host__ __device__ __forceinline__ constexpr double fANGF(int K) { return (K & 1) ? 1.0 : -1.0; }       // 1,-1,1,-1,...

// Result of one facet: slope S and how its angle A is obtained
//   code 0: A = 0;  code 1: A = AD (atan2(D2,D1), host glibc value);  code 2: A = atan2(S2,S1)
struct Facet { double S, S1, S2; int code; };

// Branch selection of VSLOPE without evaluating atan2 unless the facet direction is
// within 1e-9 (relative) of the facet diagonal; S2 = (E1-E2)/D2 is never -0, so
// atan2(S2,S1) < 0  <=>  S2 < 0.
__device__ __forceinline__ Facet vslope_dev(double E0, double E1, double E2, double D1, double D2, double DD) {
  Facet f;
  f.S1 = (E0 - E1) / D1;
  f.S2 = (E1 - E2) / D2;
  if (f.S2 < 0.) { f.S = f.S1; f.code = 0; return f; }
  bool clip;
  if (f.S1 <= 0.) clip = !(f.S1 == 0. && f.S2 == 0.);
  else {
    const double x = f.S2 * D1, y = f.S1 * D2;
    if (x > y * (1. + 1e-9)) clip = true;
    else if (x < y * (1. - 1e-9)) clip = false;
    else clip = atan2(f.S2, f.S1) > atan2(D2, D1);
  }
  if (clip) { f.S = (E0 - E2) / DD; f.code = 1; return f; }
  f.S = sqrt(f.S1 * f.S1 + f.S2 * f.S2);
  f.code = (f.S1 == 0. && f.S2 == 0.) ? 0 : 2;
  return f;
}

__device__ __forceinline__ double facet_angle(const Facet& f, double AD) {
  return f.code == 0 ? 0. : f.code == 1 ? AD : atan2(f.S2, f.S1);
}

// angle written to the raster: (float)(ANGC[K]*(PI/2) + ANGF[K]*A)   (src/dinf.cpp:367)
__device__ __forceinline__ float dinf_angle(int K, double A) { return (float)(fANGC(K) * (TD_PI / 2) + fANGF(K) * A); }

// ---- prop(): share of a D-infinity cell's flow that goes to neighbour k (src/commonLib.cpp:76-91)
__device__ __forceinline__ double aref(int i, double t) {
  // i in 0..9
  const double PI = TD_PI;
  switch (i) {
    case 0: return -t;
    case 1: return 0.;
    case 2: return t;
    case 3: return (double)(0.5 * PI);
    case 4: return PI - t;
    case 5: return (double)PI;
    case 6: return PI + t;
    case 7: return (double)(1.5 * PI);
    case 8: return 2. * PI - t;
    default: return (double)(2. * PI);
  }
}

// src/commonLib.cpp:76-91
__device__ __forceinline__ double prop_dev(float a, int k, double t) {
  const double PI = TD_PI;
  double p = 0.;
  if (k <= 0) k = k + 8;
  if (k == 1 && a > PI) a = (float)(a - 2.0 * PI);
  const double lo = aref(k - 1, t), hi = aref(k + 1, t);
  if (a > lo && a < hi) {
    const double mid = aref(k, t);
    if (a > mid) p = (hi - a) / (hi - mid);
    else p = (a - lo) / (mid - lo);
  }
  if (p < 1e-5) return -1.;
  return p;
}

// The outflow of a D-infinity cell in one step: the (at most two) receiving directions and their
// shares, identical to evaluating prop(a, k) for k = 1..8 but with one interval search and two
// divisions.  ar[0..9] is prop()'s aref[] table for the cell's row.  With j = #{i in 1..9 : ar[i] <= a}:
//   1 <= j <= 8 : a lies in [ar[j], ar[j+1])  -> directions j and j+1 (direction 1 after j = 8, reached
//                 through prop's "k == 1 && a > PI" wrap with the float-rounded a - 2*PI);
//   j == 9 (a >= 2*PI) or j == 0 (a < 0)    -> direction 1 only.
struct Outflow { int k1, k2; double p1, p2; };

// prop()'s table for one row without storing it: ArefRow{t}[i] == aref(i, t)
struct ArefRow {
  double t;
  __device__ __forceinline__ double operator[](int i) const { return aref(i, t); }
};

template <typename AR>
__device__ __forceinline__ double prop_dir1_wrapped(float a, const AR& ar) {
  const float a1 = (float)(a - 2.0 * TD_PI);
  double p = 0.;
  if (a1 > ar[0] && a1 < ar[2]) p = (a1 > ar[1]) ? (ar[2] - a1) / (ar[2] - ar[1]) : (a1 - ar[0]) / (ar[1] - ar[0]);
  return p;
}

template <typename AR>
__device__ __forceinline__ Outflow dinf_outflow_t(float a, const AR& ar) {
  Outflow o; o.k1 = o.k2 = 0; o.p1 = o.p2 = 0.;
  int j = 0;
#pragma unroll
  for (int i = 1; i <= 9; ++i) j += (a >= ar[i]) ? 1 : 0;
  double pA = 0., pB = 0.; int kA = 0, kB = 0;
  if (j >= 1 && j <= 8) {
    const double lo = ar[j - 1], mid = ar[j], hi = ar[j + 1];
    kA = j;
    pA = (a > mid) ? (hi - a) / (hi - mid) : (a - lo) / (mid - lo);
    if (j < 8) { kB = j + 1; if (a > mid) pB = (a - mid) / (hi - mid); }
    else { kB = 1; pB = prop_dir1_wrapped(a, ar); }
  } else if (j == 9) {
    kA = 1; pA = prop_dir1_wrapped(a, ar);
  } else {
    kA = 1;
    if (a > ar[0]) pA = (a > ar[1]) ? (ar[2] - a) / (ar[2] - ar[1]) : (a - ar[0]) / (ar[1] - ar[0]);
  }
  if (!(pA < 1e-5)) { o.k1 = kA; o.p1 = pA; }
  if (!(pB < 1e-5)) { if (o.k1 == 0) { o.k1 = kB; o.p1 = pB; } else { o.k2 = kB; o.p2 = pB; } }
  return o;
}
// p >= 1e-5 for p = num / den (den > 0) without the division: p is only compared with the threshold, and the rounded
// quotient can differ from num / den by half an ulp, so anything outside a 2^-40 band around 1e-5 * den is decided by
// one multiplication; inside the band (practically never) the division itself decides.
__device__ __forceinline__ bool share_counts(double num, double den) {
  const double t = 1e-5 * den;
  if (num > t * (1. + 0x1p-40)) return true;
  if (num < t * (1. - 0x1p-40)) return false;
  return !(num / den < 1e-5);
}

// Only the receivers of a cell, coded as k1 | 0x10 if there is a second one (k1 % 8 + 1) — what the dependency stencil
// needs — without evaluating the shares: same interval search as dinf_outflow, two multiplications instead of two
// divisions in the sectors 1..7; the wrap sector and angles outside [0, 2 PI) go through dinf_outflow itself.
template <typename AR>
__device__ __forceinline__ unsigned dinf_receivers(float a, const AR& ar) {
  int j = 0;
#pragma unroll
  for (int i = 1; i <= 9; ++i) j += (a >= ar[i]) ? 1 : 0;
  if (j >= 1 && j <= 7) {
    const double lo = ar[j - 1], mid = ar[j], hi = ar[j + 1];
    const bool up = a > mid;
    const bool cA = up ? share_counts(hi - a, hi - mid) : share_counts(a - lo, mid - lo);     // pA >= 1e-5
    const bool cB = up && share_counts(a - mid, hi - mid);                                      // pB >= 1e-5 (pB = 0 unless a > mid)
    if (cA) return (unsigned)j | (cB ? 0x10u : 0u);
    return cB ? (unsigned)(j + 1) : 0u;
  }
  const Outflow o = dinf_outflow_t(a, ar);
  return (unsigned)o.k1 | (o.k2 ? 0x10u : 0u);
}

// prop()'s table of a row with the sector widths and their reciprocals (strips whose rows all have the same cell size
// keep ONE of these; the sweeps then get every share with a division by a table constant, rowfact.cuh's div_const)
struct PropRow {
  double ar[10];      // aref[]
  double den[9];      // ar[j + 1] - ar[j]
  double rden[9];     // RN(1 / den[j])
  int safe;           // every den satisfies the precondition of the reciprocal division
  int uniform;        // the strip has one table (else: per-row angles, plain divisions)
};

// host: the table for row angle t = atan2(dy, dx) (src/commonLib.cpp:78 builds aref[] with these expressions)
// The per-row angle table of a strip: theta[0 .. ny) = atan2(dy, dx) of the owned rows, [ny .. 2 ny) = atan2(dx, dy),
// theta[2 ny] / theta[2 ny + 1] = atan2(dy, dx) of the rows above / below the strip (the neighbour strips' edge rows: the
// reference evaluates a contributor with ITS row's cell sizes, getdxdyc(jn), src/areadinf.cpp:199-201).  row = strip row 0 .. ny + 1.
__host__ __device__ __forceinline__ double theta_of_row(const double* theta, int ny, int row) {
  return row < 1 ? theta[2 * ny] : (row > ny ? theta[2 * ny + 1] : theta[row - 1]);
}

inline void make_prop_row(double t, bool uniform, PropRow* P) {
  const double PI = TD_PI;
  const double ar[10] = {-t, 0., t, (double)(0.5 * PI), PI - t, (double)PI, PI + t, (double)(1.5 * PI), 2. * PI - t, (double)(2. * PI)};
  P->safe = 1; P->uniform = uniform ? 1 : 0;
  for (int i = 0; i < 10; i++) P->ar[i] = ar[i];
  for (int i = 0; i < 9; i++) {
    P->den[i] = ar[i + 1] - ar[i];
    P->rden[i] = 1. / P->den[i];
    unsigned long long b; memcpy(&b, &P->den[i], 8);
    const unsigned long long ex = (b >> 52) & 0x7ffull, mant = b & 0xfffffffffffffull;
    if ((b >> 63) != 0 || ex <= 64 || ex >= 1983 || mant == 0xfffffffffffffull) P->safe = 0;
  }
}


// The receivers of a cell as the byte the dependency stencils work with: bits 0-3 = first receiver k1 (0 = none), 0x10 = there
// is a second one (always k1 % 8 + 1), 0x40 = "upper": the angle lies in sector k1 - 1 and the sector's lower direction
// got a share below 1e-5 (dropped), so the only receiver is the sector's upper direction, 0x80 = irregular: angles outside
// [0, 2 PI), a wrap-sector cell whose only receiver is direction 1, an upper cell with k1 = 8 — the shares of such cells come
// from the interval search.  Regular cells get their shares from the sector table alone:
//   sector j = k1 (not upper) or k1 - 1 (upper);  first receiver of a non-upper cell: (ar[j+1] - a) / den[j];
//   second receiver, or the only receiver of an upper cell: (a - ar[j]) / den[j] — except in the wrap sector j = 8, whose
//   second receiver is direction 1 reached through prop()'s float-rounded a - 2 PI: ((float)(a - 2 PI) - ar[0]) / den[0].
template <typename AR>
__device__ __forceinline__ unsigned dinf_node_code(float a, const AR& ar) {
  int j = 0;
#pragma unroll
  for (int i = 1; i <= 9; ++i) j += (a >= ar[i]) ? 1 : 0;
  if (j >= 1 && j <= 7) {
    const double lo = ar[j - 1], mid = ar[j], hi = ar[j + 1];
    const bool up = a > mid;
    const bool cA = up ? share_counts(hi - a, hi - mid) : share_counts(a - lo, mid - lo);     // pA >= 1e-5
    const bool cB = up && share_counts(a - mid, hi - mid);                                      // pB >= 1e-5 (pB = 0 unless a > mid)
    if (cA) return (unsigned)j | (cB ? 0x10u : 0u);
    if (!cB) return 0u;
    return j < 7 ? ((unsigned)(j + 1) | 0x40u) : (8u | 0x80u);
  }
  const Outflow o = dinf_outflow_t(a, ar);
  if (o.k1 == 0) return 0u;
  if (j == 8 && o.k1 == 8) return 8u | (o.k2 ? 0x10u : 0u);      // wrap sector with direction 8 counted: regular
  return (unsigned)o.k1 | (o.k2 ? 0x10u : 0u) | 0x80u;
}
// The 4-bit receiver field of a D-infinity node word (bits 8-11) and what the word's 0x2000 bit means with it:
//   0 none; 1..8 regular, k1 = f, 0x2000 = there is a second receiver; 10..15 upper, k1 = f - 8, single receiver;
//   9 irregular, single receiver: k1 = 1, or k1 = 8 if 0x2000 is set (the bit has no other use there).
__host__ __device__ __forceinline__ unsigned dinf_node_bits(unsigned code) {      // (field << 8) | 0x2000 flag, from dinf_node_code's byte
  const unsigned k = code & 0xfu;
  if (code & 0x80u) return (9u << 8) | (k == 8u ? 0x2000u : 0u);
  if (code & 0x40u) return (k + 8u) << 8;
  return (k << 8) | ((code & 0x10u) ? 0x2000u : 0u);
}
__host__ __device__ __forceinline__ int dinf_node_k1(unsigned nd) {
  const unsigned f = (nd >> 8) & 0xfu;
  return (int)(f <= 8u ? f : (f == 9u ? ((nd & 0x2000u) ? 8u : 1u) : f - 8u));
}
__host__ __device__ __forceinline__ int dinf_node_k2(unsigned nd) {       // the second receiver, 0 = none
  const unsigned f = (nd >> 8) & 0xfu;
  return (f >= 1u && f <= 8u && (nd & 0x2000u)) ? (int)(f & 7u) + 1 : 0;
}

// RN(x / d) for a divisor whose correctly rounded reciprocal y = RN(1 / d) is known (see rowfact.cuh's div_const: Markstein's
// two corrections with exact residuals; PropRow::safe / RowFact::safe state the precondition)
__device__ __forceinline__ double div_recip(double x, double d, double y) {
  const double q0 = x * y;
  const double q1 = fma(fma(-d, q0, x), y, q0);
  return fma(fma(-d, q1, x), y, q1);
}

__device__ __forceinline__ Outflow dinf_outflow(float a, const double* ar) { return dinf_outflow_t(a, ar); }
__device__ __forceinline__ Outflow dinf_outflow(float a, double t) { return dinf_outflow_t(a, ArefRow{t}); }

}  // namespace td
