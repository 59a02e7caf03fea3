// Per-row constants of the flow-direction stencils (rowfact.cuh).
#include "rowfact.cuh"

namespace td {
namespace {
__device__ __forceinline__ bool recip_safe(double d) {
  // finite, normal, positive, significand not all ones (precondition of div_const)
  const unsigned long long b = (unsigned long long)__double_as_longlong(d);
  const unsigned long long ex = (b >> 52) & 0x7ffull, mant = b & 0xfffffffffffffull;
  return (b >> 63) == 0 && ex > 64 && ex < 1983 && mant != 0xfffffffffffffull;
}
__global__ void k_row_factors(const double* __restrict__ dxc, const double* __restrict__ dyc, const double* __restrict__ thA,
                              const double* __restrict__ thB, RowFact* __restrict__ out, int ny) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ny) return;
  RowFact f;
  const double dx = dxc[j], dy = dyc[j];
  f.dx = dx; f.dy = dy; f.dd = sqrt(dx * dx + dy * dy);
  f.rdx = 1. / f.dx; f.rdy = 1. / f.dy; f.rdd = 1. / f.dd;
  f.fE = 1. / sqrt(dx * dx); f.fN = 1. / sqrt(dy * dy); f.fD = 1. / sqrt(dx * dx + dy * dy);   // src/d8.cpp:369-377
  f.adA = thA ? thA[j] : 0.; f.adB = thB ? thB[j] : 0.;
  f.dxf = (float)dx; f.dyf = (float)dy; f.rdxf = 1.0f / f.dxf; f.rdyf = 1.0f / f.dyf; f.rddf = 1.0f / (float)f.dd;
  f.fEf = (float)f.fE; f.fNf = (float)f.fN; f.fDf = (float)f.fD;
  f.safe = (recip_safe(f.dx) && recip_safe(f.dy) && recip_safe(f.dd)) ? 1 : 0;
  f.pad = 0;
  out[j] = f;
}
}  // namespace

void launch_row_factors(const double* dxc, const double* dyc, const double* thA, const double* thB, RowFact* out, int ny, cudaStream_t st) {
  k_row_factors<<<(ny + 127) / 128, 128, 0, st>>>(dxc, dyc, thA, thB, out, ny);
  TD_LAUNCHED();
}

}  // namespace td
