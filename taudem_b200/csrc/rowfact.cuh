// Per-row constants of the flow-direction stencils (one entry per strip row, built by k_row_factors once per call):
// cell sizes, their reciprocals for the correctly rounded "division by a row constant" below, the D8 distance factors
// 1/sqrt((d1 dx)^2 + (d2 dy)^2) (src/d8.cpp:369-377), float copies for the pre-screens, the facet diagonal angles.
#pragma once
#include "common.cuh"

namespace td {

struct RowFact {
  double dx, dy, dd;          // cell size, diagonal sqrt(dx^2 + dy^2)
  double rdx, rdy, rdd;       // RN(1 / dx), RN(1 / dy), RN(1 / dd)
  double fE, fN, fD;          // 1 / sqrt(dx dx), 1 / sqrt(dy dy), 1 / sqrt(dx dx + dy dy)
  double adA, adB;            // atan2(dy, dx), atan2(dx, dy) (host glibc values)
  float dxf, dyf, rdxf, rdyf, rddf, fEf, fNf, fDf;
  int safe;                   // the reciprocals satisfy the precondition of div_const() for all three divisors
  int pad;
};

// RN(x / d) for a divisor whose correctly rounded reciprocal y = RN(1 / d) is known: two Newton corrections of
// q = x y with exact residuals (Markstein, "Computation of elementary functions on the IBM RISC System/6000 processor",
// 1990: if y = RN(1 / d) and q is a faithful quotient, RN(q + (x - d q) y) is the correctly rounded quotient unless
// the significand of d is all ones — RowFact::safe excludes that case and non-finite / subnormal divisors).  The
// operands here are elevation differences and cell sizes: no overflow, and x / d is never subnormal in double.
__device__ __forceinline__ double div_const(double x, double d, double y) {
  const double q0 = x * y;
  const double q1 = fma(fma(-d, q0, x), y, q0);
  return fma(fma(-d, q1, x), y, q1);
}

void launch_row_factors(const double* dxc, const double* dyc, const double* thA, const double* thB, RowFact* out, int ny, cudaStream_t st);

}  // namespace td
